// LZ77 parameter explorer (developer tool): hash-chain matcher variants, cost via Huffman code lengths.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <math.h>
#define WIN 32768
static uint8_t *buf; static size_t n;
static int HB_A=12, NB_A=4, DEPTH_A=32, HB_B=12, NB_B=0, DEPTH_B=1, LAZY=32, NICE=258, SUB=0, CHUNK=262144, MINM=4, GOODCUT=0, RACY=0, CS=0, TWO=0, STRIDE=1;
static inline uint64_t ld8(const uint8_t*p){uint64_t v; memcpy(&v,p,8); return v;}
static inline uint32_t hashn(const uint8_t*p,int nb,int hb){ uint64_t v=ld8(p); if(nb<8) v&=((1ull<<(8*nb))-1); return (uint32_t)((v*0x9E3779B185EBCA87ull)>>(64-hb)); }
static int mlen(const uint8_t*a,const uint8_t*b,int maxl){int l=0; while(l<maxl&&a[l]==b[l])l++; return l;}
static void huff_len(const uint32_t*f,int nsym,int *len){ // plain Huffman, unlimited length (approximation)
  int idx[600]; uint64_t w[600]; int par[600]; int m=0; for(int i=0;i<nsym;i++){len[i]=0; if(f[i]){idx[m]=i;w[m]=f[i];m++;}}
  if(m==0)return; if(m==1){len[idx[0]]=1;return;}
  int tot=m; int alive[600]; int na=m; for(int i=0;i<m;i++)alive[i]=i;
  while(na>1){ int a=0,b=1; if(w[alive[b]]<w[alive[a]]){a=1;b=0;} for(int i=2;i<na;i++){ if(w[alive[i]]<w[alive[a]]){b=a;a=i;} else if(w[alive[i]]<w[alive[b]]) b=i; }
    w[tot]=w[alive[a]]+w[alive[b]]; par[alive[a]]=tot; par[alive[b]]=tot; int hi=a>b?a:b, lo=a<b?a:b; alive[lo]=tot; alive[hi]=alive[na-1]; na--; tot++; }
  for(int i=0;i<m;i++){int d=0,x=i; while(x!=tot-1){x=par[x];d++;} len[idx[i]]=d>15?15:d;}
}
static int lcode(int len){int l=len-3; if(l<8)return l; if(l==255)return 28; int nb=31-__builtin_clz(l); return 4*(nb-1)+((l>>(nb-2))&3);}
static int lext(int c){return c<8||c==28?0:(c-4)/4;}
static int dcode(int dist){int d=dist-1; if(d<4)return d; int nb=31-__builtin_clz(d); return 2*nb+((d>>(nb-1))&1);}
static int dext(int c){return c<4?0:(c-2)/2;}
int main(int argc,char**argv){
  const char*fn=argv[1]; size_t limit=0;
  for(int i=2;i<argc;i++){ char*a=argv[i]; int v=atoi(strchr(a,'=')?strchr(a,'=')+1:"0");
    if(!strncmp(a,"ha=",3))HB_A=v; else if(!strncmp(a,"na=",3))NB_A=v; else if(!strncmp(a,"da=",3))DEPTH_A=v; else if(!strncmp(a,"hb=",3))HB_B=v; else if(!strncmp(a,"nb=",3))NB_B=v; else if(!strncmp(a,"db=",3))DEPTH_B=v;
    else if(!strncmp(a,"lazy=",5))LAZY=v; else if(!strncmp(a,"nice=",5))NICE=v; else if(!strncmp(a,"sub=",4))SUB=v; else if(!strncmp(a,"chunk=",6))CHUNK=v; else if(!strncmp(a,"min=",4))MINM=v; else if(!strncmp(a,"limit=",6))limit=(size_t)v<<20; else if(!strncmp(a,"good=",5))GOODCUT=v; else if(!strncmp(a,"racy=",5))RACY=v; else if(!strncmp(a,"cs=",3))CS=v; else if(!strncmp(a,"two=",4))TWO=v; else if(!strncmp(a,"stride=",7))STRIDE=v; }
  FILE*f=fopen(fn,"rb"); fseek(f,0,SEEK_END); n=ftell(f); fseek(f,0,SEEK_SET); if(limit&&n>limit)n=limit; buf=malloc(n+16); if(fread(buf,1,n,f)!=n)return 1; memset(buf+n,0,16);
  int *headA=malloc(sizeof(int)<<HB_A), *headB=malloc(sizeof(int)<<(NB_B?HB_B:1)); int *prevA=malloc(sizeof(int)*(n+1)), *prevB=NB_B?malloc(sizeof(int)*(n+1)):0;
  int *blen=malloc(sizeof(int)*(CHUNK+1)), *bdist=malloc(sizeof(int)*(CHUNK+1));
  double total_bits=0; uint64_t ntok=0, hops=0, cmps=0, nsearch=0;
  for(size_t c0=0;c0<n;c0+=CHUNK){ size_t c1=c0+CHUNK<n?c0+CHUNK:n; size_t h0=c0>=WIN?c0-WIN:0;
    for(int i=0;i<(1<<HB_A);i++)headA[i]=-1; if(NB_B)for(int i=0;i<(1<<HB_B);i++)headB[i]=-1;
    // build chains for [h0,c1)
    if(RACY){ for(size_t b0=h0;b0<c1;b0+=RACY){ size_t b1=b0+RACY<c1?b0+RACY:c1; for(size_t p=b0;p<b1;p++){ if(p+NB_A<=n){uint32_t h=hashn(buf+p,NB_A,HB_A); int o=headA[h]; prevA[p]=(o>=(int)b0)?-2:o;} else prevA[p]=-1; }
        /* positions whose bucket was already overwritten inside this batch read the pre-batch head: emulate by remembering it */
        for(size_t p=b0;p<b1;p++){ if(p+NB_A<=n){uint32_t h=hashn(buf+p,NB_A,HB_A); if(prevA[p]==-2){ int q=headA[h]; while(q>=(int)b0) q=prevA[q]; prevA[p]=q; } headA[h]=(int)p; } } } }
    else for(size_t p=h0;p<c1;p++){ if(p+NB_A<=n){uint32_t h=hashn(buf+p,NB_A,HB_A); prevA[p]=headA[h]; headA[h]=(int)p;} else prevA[p]=-1;
      if(NB_B){ if(p+NB_B<=n){uint32_t h=hashn(buf+p,NB_B,HB_B); prevB[p]=headB[h]; headB[h]=(int)p;} else prevB[p]=-1; } }
    // per-position best
    static unsigned char *mark=0; if(!mark) mark=malloc(CHUNK+2);
    int npass = TWO?2:1;
    for(int pass=0; pass<npass; pass++){
    int DEPTH_SAVE=DEPTH_A; if(TWO && pass==0) DEPTH_A=TWO;
    for(size_t p=c0;p<c1;p++){ int maxl=258; if(TWO && pass==1 && !mark[p-c0]) continue;
      if(TWO && pass==0 && STRIDE==2 && ((p-c0)&1)){ blen[p-c0]=-1; continue; }   /* stride 2: odd positions are derived from their neighbours below */ size_t lim=c1; if(SUB){ size_t sb=c0+((p-c0)/SUB+1)*SUB; if(sb<lim)lim=sb; } if(p+maxl>lim)maxl=lim-p;
      int bl=MINM-1,bd=0; nsearch++;
      int depth=DEPTH_A;
      if(CS){ int off=0, seen=0; int node=prevA[p]; for(int k=0;k<depth;k++){ if(node<0)break; int q=node-off; if(q<(int)h0||(int)p-q>WIN)break; hops++; int dist=p-q;
          if(dist<=seen){node=prevA[node];continue;} seen=dist;
          if(buf[q+bl]!=buf[p+bl]&&bl<maxl){node=prevA[node];continue;} cmps++; int l=mlen(buf+p,buf+q,maxl);
          if(l>bl){bl=l;bd=dist; if(l>=NICE||l>=maxl)break; if(l>=CS && p+l-4+NB_A<=n && p+l-4<c1){ off=l-4; node=prevA[p+off]; seen=0; /* restart on the selective chain; nearer ones re-skipped below */ seen=0; continue;} }
          node=prevA[node]; }
        if(bl<MINM||bl>maxl){bl=0;} blen[p-c0]=bl; bdist[p-c0]=bd; continue; }
      for(int q=prevA[p],k=0;q>=(int)h0&&k<depth&&(int)p-q<=WIN;q=prevA[q],k++){ hops++; if(buf[q+bl]!=buf[p+bl]&&bl<maxl)continue; cmps++; int l=mlen(buf+p,buf+q,maxl); if(l>bl){bl=l;bd=p-q; if(l>=NICE||l>=maxl)break; if(GOODCUT&&l>=GOODCUT&&depth>k+1+DEPTH_A/4)depth=k+1+DEPTH_A/4;} }
      if(NB_B&&bl<NICE&&bl<maxl){ for(int q=prevB[p],k=0;q>=(int)h0&&k<DEPTH_B&&(int)p-q<=WIN;q=prevB[q],k++){ hops++; if(buf[q+bl]!=buf[p+bl]&&bl<maxl)continue; cmps++; int l=mlen(buf+p,buf+q,maxl); if(l>bl){bl=l;bd=p-q; if(l>=NICE||l>=maxl)break;} } }
      if(bl<MINM||bl>maxl){bl=0;} blen[p-c0]=bl; bdist[p-c0]=bd; }
    DEPTH_A=DEPTH_SAVE;
    if(TWO && pass==0 && STRIDE==2){ for(size_t p=c0+1;p<c1;p+=2){ int bl=0,bd=0; int l0=blen[p-1-c0]; if(l0-1>=MINM){bl=l0-1;bd=bdist[p-1-c0];}
        if(p+1<c1){ int l1=blen[p+1-c0], d1=bdist[p+1-c0]; if(l1>0 && (int)(p-c0)>=0 && p>=(size_t)d1+h0 && buf[p]==buf[p-d1] && l1+1>bl && l1+1<=258){bl=l1+1;bd=d1;} }
        blen[p-c0]=bl; bdist[p-c0]=bd; } }
    if(TWO && pass==0){ memset(mark,0,CHUNK+2); size_t p=c0; while(p<c1){ int l=blen[p-c0]; mark[p-c0]=1; if(p+1<c1) mark[p+1-c0]=1; if(l&&LAZY&&l<LAZY&&p+1<c1&&blen[p+1-c0]>l) l=0; p+= l?l:1; } }
    }
    // parse greedy/lazy(1)
    uint32_t lf[286]={0},df[30]={0}; double xbits=0; size_t p=c0;
    while(p<c1){ int l=blen[p-c0]; if(l&&LAZY&&l<LAZY&&p+1<c1&&blen[p+1-c0]>l) l=0;
      if(l){ int lc=lcode(l),dc=dcode(bdist[p-c0]); lf[257+lc]++; df[dc]++; xbits+=lext(lc)+dext(dc); p+=l; } else { lf[buf[p]]++; p++; } ntok++; }
    lf[256]=1; int ll[286],dl[30]; huff_len(lf,286,ll); huff_len(df,30,dl); double bits=xbits; for(int i=0;i<286;i++)bits+=(double)lf[i]*ll[i]; for(int i=0;i<30;i++)bits+=(double)df[i]*dl[i];
    bits+=70*8; double stored=(c1-c0)*8.0+40; if(stored<bits)bits=stored; total_bits+=bits+35; }
  printf("%-28s ratio %.3f  tokens %.2fM (%.1f B/tok) hops/pos %.1f cmps/pos %.2f\n", argv[2]?argv[2]:"", n*8.0/total_bits, ntok/1e6, (double)n/ntok, (double)hops/nsearch, (double)cmps/nsearch);
  return 0; }
